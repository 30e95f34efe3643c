#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""
bench.py -- headline benchmark of the per-pixel iteration hot path.

    python bench.py --gpus N --steps K --warmup W [--workload config2]
    python bench.py --impl reference ...      (CPU arm: oracle port, host cores)

A "step" is one pass of the hot path over one whole frame (all tiles of the
image, 8 294 400 points at 4K).  Metric (BASELINE.json): effective pixel-
iterations per second = sum over pixels of stop_iter / time, BLA-skipped
iterations counted, in Gpix-iter/s; s/frame is ms_per_step / 1000.

  value   device-resident: c_pix and the output planes live in HBM, timed with
          CUDA events on the launching stream (fsb_stats.kernel_ms), L2 flushed
          between steps.
  e2e     the reference-facing seam `numba_cycle_call` (-> fsb_frame_run) with
          pinned HOST buffers: H2D of c_pix, kernel, D2H of Z/U/stop_* inside
          the timed region (wall clock around the call).

N > 1 (torchrun): every rank renders its own full frame (frames of a zoom
movie are independent: weak scaling, no data-path collective);
torch.distributed is used only for the barrier and the max-over-ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

from fractalshades_b200.views import VIEWS  # noqa: E402

_DJ = VIEWS["deep_julia_2608"]
_BS = VIEWS["bs_deep_julia_2430"]
_STD = dict(M_divergence=1e3, epsilon_stationnary=1e-3)
HEADLINE = "config3"     # BASELINE.json's target: the 1e-1000 4K frame

WORKLOADS = {
    # BASELINE.json configs[0]
    "config1": dict(
        name="Mandelbrot calc_std_div 800x800 max_iter 5000 (x=-1, y=0, dx=5)",
        kind="std_M2", x=-1.0, y=0.0, dx=5.0, nx=800, xy_ratio=1.0,
        calc=dict(max_iter=5000, M_divergence=1000., epsilon_stationnary=1e-3)),
    # BASELINE.json configs[1] -- the bench default
    "config2": dict(
        name="Perturbation_mandelbrot dx=1e-250 3840x2160 max_iter 1e6 BLA 1e-6 dzndc",
        kind="perturb_M2", precision=270, x=_DJ["x"][:290], y=_DJ["y"][:290],
        dx="1e-250", nx=3840, xy_ratio=16 / 9.,
        calc=dict(max_iter=1000000, BLA_eps=1e-6, interior_detect=False,
                  calc_dzndc=True, **_STD)),
    # BASELINE.json configs[2]
    "config3": dict(
        name="Perturbation_mandelbrot dx=1e-1000 (Xrange) 3840x2160 max_iter 1e7 BLA+rebasing",
        kind="perturb_M2", precision=1020, x=_DJ["x"][:1040], y=_DJ["y"][:1040],
        dx="1e-1000", nx=3840, xy_ratio=16 / 9.,
        calc=dict(max_iter=10000000, BLA_eps=1e-6, interior_detect=False,
                  calc_dzndc=True, **_STD)),
    # BASELINE.json configs[3]
    "config4": dict(
        name="Perturbation_burning_ship dx=1e-500 hessian 3840x2133",
        kind="perturb_BS", init=dict(flavor="Burning ship"), precision=520,
        x=_BS["x"][:540], y=_BS["y"][:540], dx="1e-500", nx=3840, xy_ratio=1.8,
        theta_deg=12.0, skew=_BS["skew"],
        calc=dict(max_iter=500000, M_divergence=1e3, BLA_eps=1e-6,
                  calc_hessian=True)),
}


# ---------------------------------------------------------------------------
def make_fractal(w, nx=None):
    import fractalshades_b200.models as fsm
    from fractalshades_b200 import settings
    settings.no_newton = True      # reference point = image centre
    cls = {"std_M2": fsm.Mandelbrot, "std_BS": fsm.Burning_ship,
           "perturb_M2": fsm.Perturbation_mandelbrot,
           "perturb_BS": fsm.Perturbation_burning_ship}[w["kind"]]
    f = cls(tempfile.mkdtemp(prefix="fsb_bench_"), **w.get("init", {}))
    zoom = dict(x=w["x"], y=w["y"], dx=w["dx"], nx=nx or w["nx"],
                xy_ratio=w["xy_ratio"], theta_deg=w.get("theta_deg", 0.),
                **w.get("skew", {}))
    if w["kind"].startswith("perturb"):
        zoom["precision"] = w["precision"]
    f.zoom(**zoom)
    return f


def frame_c_pix(f, tiles=None, shapes=None):
    """ tile-ordered pixel offsets (the scheduler's point list); `shapes`
    collects the (width, height) of each tile """
    out = []
    for cs in (tiles if tiles is not None else f.chunk_slices()):
        pos = f.chunk_pixel_pos(cs, False, None)
        if shapes is not None:
            shapes.append((pos.shape[1], pos.shape[0]))
        out.append(np.ravel(pos))
    return np.ascontiguousarray(np.concatenate(out))


class ClockSampler:
    """ nvidia-smi clocks / throttle reasons during the timed region """
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.device), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                inside = (t0 - 0.05 <= ts <= t1 + 0.15)
                if inside:
                    sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                if inside:
                    for nme, val in zip(names, parts[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(nme)
            except ValueError:
                continue
        if not sm:      # timed region shorter than the sampling period
            sm = [float(p.split(",")[1]) for _, p in self.lines[-3:] if len(p.split(",")) > 2]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bench_config(w, f, world):
    """ `config` of the JSON line: the same keys and values in both arms """
    return {"workload": w["name"], "npts_per_gpu": int(f.nx) * int(f.ny),
            "frames_per_step": int(world), "tile": 200}


def flops_per_unit(kind, w):
    """ Algorithmic FP64 flops (FMA = 2) per executed iteration / BLA step,
    counted from the reference source. """
    if kind == "std_M2":
        return 31., 0.
    if kind == "perturb_M2":
        c = w["calc"]
        it = 17. + (18. if c.get("calc_dzndc") else 0.) + (23. if c.get("interior_detect") else 0.)
        bla = 14. + (6. if c.get("calc_dzndc") else 0.) + (6. if c.get("interior_detect") else 0.)
        return it, bla
    if kind == "perturb_BS":
        c = w["calc"]
        it = 23. + (64. if c.get("calc_hessian") else 0.)
        bla = 14. + (12. if c.get("calc_hessian") else 0.)
        return it, bla
    return 40., 0.


# ---------------------------------------------------------------------------
def oracle_frame(w, f, spec):
    """ tables dict for the CPU oracle (test infrastructure, cpu_baseline only) """
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import parity_common as pc
    if w["kind"].startswith("std"):
        return None
    t = f.frame_tables()
    t0 = time.time()
    pc.oracle_fill_tables(t)
    return t, time.time() - t0


def oracle_run(w, f, t, c_pix, nthreads=0):
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import oracle_lib as ol
    import fractalshades_b200.models as fsm
    if w["kind"] == "std_M2":
        return ol.std_m2(c_pix, complex(f.x, f.y), float(f.dx), f.lin_mat,
                         nthreads=nthreads, **w["calc"])[3]
    if w["kind"] == "std_BS":
        return ol.std_bs(fsm.get_flavor_int(f.flavor), c_pix, complex(f.x, f.y),
                         float(f.dx), f.lin_mat, nthreads=nthreads, **w["calc"])[3]
    return ol.perturb(t, c_pix, nthreads)[3]


def bind_spec(f, w):
    kw = dict(calc_name="bench", subset=None, **w["calc"])
    ret = type(f).calc_std_div.__wrapped__(f, **kw)
    for k, v in kw.items():
        setattr(f, k, v)
    ret["set_state"]()(f)
    spec = ret["iterate"]()
    f._kernel_options = vars(spec).copy()
    return spec


def host_threads():
    """ all usable host cores (torchrun exports OMP_NUM_THREADS=1: the thread
    count is therefore passed explicitly to the OpenMP port) """
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample(w, f, target_s=12.0, nthreads=0):
    """ Time the oracle port on a bounded 1-in-k sample of the frame's tiles. """
    nthreads = nthreads or host_threads()
    spec = bind_spec(f, w)
    t = None
    t_tables = 0.
    if w["kind"].startswith("perturb"):
        t, t_tables = oracle_frame(w, f, spec)
    tiles = list(f.chunk_slices())
    # calibrate on one central tile
    mid = tiles[len(tiles) // 2]
    c0 = frame_c_pix(f, [mid])
    t0 = time.time()
    si = oracle_run(w, f, t, c0, nthreads)
    dt = max(time.time() - t0, 1e-4)
    n_tiles = int(max(1, min(len(tiles), target_s / dt)))
    step = max(1, len(tiles) // n_tiles)
    sample = tiles[::step][:n_tiles]
    c = frame_c_pix(f, sample)
    t0 = time.time()
    si = oracle_run(w, f, t, c, nthreads)
    dt = time.time() - t0
    return {"iters": int(si.sum(dtype=np.int64)), "seconds": dt,
            "n_tiles": len(sample), "n_tiles_total": len(tiles),
            "npts": int(c.shape[0]), "tables_s": t_tables, "c_pix": c, "tables": t}


def run_reference_arm(args, w, rank, world):
    """ --impl reference: the reference's CPU implementation of the path.  The
    reference is Python/numba and cannot travel to the GPU box, so this arm
    times the oracle port (oracle/, C++/OpenMP restatement pinned bit-exact
    against the reference) with all host threads, on a bounded sample. """
    if rank != 0:
        return
    f = make_fractal(w, args.nx)
    cores = host_threads()
    s = cpu_sample(w, f, target_s=args.cpu_seconds, nthreads=cores)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        si = oracle_run(w, f, s["tables"], s["c_pix"], cores)
        if i >= args.warmup:
            times.append(time.time() - t0)
    iters = int(si.sum(dtype=np.int64))
    ms = 1e3 * float(np.mean(times)) if times else 1e3 * s["seconds"]
    val = iters / (ms * 1e-3) / 1e9
    sample = (f"{s['n_tiles']} of {s['n_tiles_total']} tiles (1-in-"
              f"{max(1, s['n_tiles_total'] // s['n_tiles'])}), {s['npts']} px")
    line = {
        "impl": "reference", "metric": "effective pixel-iterations per second",
        "value": val, "unit": "Gpix-iter/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": bench_config(w, f, args.gpus),
        "build": "oracle/fs_oracle.cpp (C++/OpenMP port of the reference's numba loops)",
        "cpu_baseline": {"value": val, "unit": "Gpix-iter/s", "cores": cores,
                         "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gpix-iter/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "s_per_frame_extrapolated": ms * 1e-3 * s["n_tiles_total"] / s["n_tiles"],
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
PP_FIELDS = {"config2": ("cont_iter", "DEM"), "config3": ("cont_iter", "DEM"),
             "config4": ("cont_iter", "normal")}


class Reducer:
    """ max / sum over ranks on the device (torch.distributed is plumbing only) """

    def __init__(self, dist):
        self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _red(self, v, op):
        if self.dist is None:
            return float(v)
        import torch
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t[0])

    def max(self, v):
        return self._red(v, self.dist.ReduceOp.MAX if self.dist else None)

    def sum(self, v):
        return self._red(v, self.dist.ReduceOp.SUM if self.dist else None)


def run_workload(args, wname, lib, rank, local_rank, world, dist, with_cpu):
    """ One workload on this rank's GPU:
      value     device-resident full-frame launches (CUDA events, L2 flushed)
      e2e       the public call a user makes for the frame's fields --
                postproc.frame_fields (pixel kernels + fused post-processing,
                fsb_frame_run_grid_pp) for the perturbation models, the raw seam
                with per-tile axes for the standard ones -- host buffers, wall clock
      e2e_raw   the reference-facing seam numba_cycle_call with the raw planes
                (Z, U, stop_reason, stop_iter) coming back to pinned host memory
      strong    one frame's tiles dealt to the N ranks (multi.tiles_for_rank),
                every rank runs the public call on its share: s/frame vs N
    Returns the record on rank 0, None elsewhere. """
    from fractalshades_b200 import _native, multi, postproc
    from fractalshades_b200.core import TileAxes, tile_shape_arrays
    red = Reducer(dist)
    w = WORKLOADS[wname]
    line = None
    # ---- per-frame setup (reference orbit, dZndc path, BLA tree) ----
    t0 = time.time()
    f = make_fractal(w, args.nx)
    f.calc_std_div(calc_name="bench", subset=None, **w["calc"])
    setup_s = time.time() - t0
    indep = f._calc_data["bench"]["cycle_indep_args"]
    perturb = (indep[0] == "perturb")
    frame = indep[1] if perturb else None
    setup_ms = frame.setup_ms() if perturb else {}

    state = f._calc_data["bench"]["state"]
    n_Z, n_U = len(state.codes[0]), len(state.codes[1])
    zdt = np.dtype(state.complex_type)

    all_tiles = list(f.chunk_slices())
    axes = TileAxes(f, all_tiles)
    npts = axes.npts
    tile_w, tile_h = axes.tw, axes.th
    n_tiles = int(tile_w.shape[0])
    Z = _native.pinned_empty((n_Z, npts), zdt)
    U = _native.pinned_empty((max(n_U, 1), npts), np.int32)
    sr = _native.pinned_empty((1, npts), np.int8)
    si = _native.pinned_empty((1, npts), np.int32)

    # ---- device-resident buffers ----
    def dalloc(nbytes):
        p = lib.fsb_dev_alloc(int(nbytes))
        if not p:
            raise RuntimeError(lib.fsb_last_error().decode())
        return p
    d_c = dalloc(npts * 16)
    d_Z = dalloc(n_Z * npts * zdt.itemsize)
    d_U = dalloc(max(n_U, 1) * npts * 4)
    d_sr = dalloc(npts)
    d_si = dalloc(npts * 4)
    c_host = frame_c_pix(f)
    _native.check(lib, lib.fsb_memcpy_h2d(d_c, _native.ptr(c_host), npts * 16))
    del c_host

    stats = _native.FsbStats()

    def step_device():
        if perturb:
            rc = lib.fsb_frame_run_tiles_device(frame.ptr, n_tiles, _native.ptr(tile_w),
                                                _native.ptr(tile_h), d_c, d_Z, d_U, d_sr,
                                                d_si, stats)
        else:
            rc = lib.fsb_std_run_tiles_device(indep[1], n_tiles, _native.ptr(tile_w),
                                              _native.ptr(tile_h), d_c, d_Z, d_sr, d_si,
                                              stats)
        _native.check(lib, rc)
        return stats.kernel_ms

    def step_raw():
        t0 = time.perf_counter()
        rc = f.numba_cycle_call((axes, Z, U[:n_U], sr, si), indep)
        assert rc == 0
        return (time.perf_counter() - t0) * 1e3

    pp_fields = PP_FIELDS.get(wname) if perturb else None
    pp_out = {}

    def step_api(tiles=None):
        """ the public call; returns (ms, sum of stop_iter of the points done) """
        t0 = time.perf_counter()
        if pp_fields:
            out, st = postproc.frame_fields(f, "bench", fields=pp_fields, copy=False,
                                            tiles=tiles)
            pp_out.update(out)
            n_it = int(st["sum_stop_iter"])
        else:
            assert tiles is None
            rc = f.numba_cycle_call((axes, Z, U[:n_U], sr, si), indep)
            assert rc == 0
            n_it = int(type(f)._last_stats["sum_stop_iter"])
        return (time.perf_counter() - t0) * 1e3, n_it

    # ---- warm-up ----
    for _ in range(max(args.warmup, 1)):
        step_device()
    step_raw()
    step_api()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)

    # ---- timed: device-resident ----
    red.barrier()
    t_begin = time.time()
    dev_ms = []
    for _ in range(args.steps):
        _native.check(lib, lib.fsb_flush_l2())      # untimed: cold L2 per step
        dev_ms.append(step_device())
    red.barrier()
    kstats = stats.as_dict()
    sum_iter = int(kstats["sum_stop_iter"])

    # ---- timed: end to end, the public call (host buffers) ----
    red.barrier()
    api_ms = []
    for _ in range(args.steps):
        ms, n_it = step_api()
        api_ms.append(ms)
    red.barrier()
    assert n_it == sum_iter, "e2e / device mismatch"
    if pp_fields:
        assert int(np.count_nonzero(pp_out["stop_reason"] >= 0)) == npts

    # ---- timed: end to end through the seam, raw planes ----
    red.barrier()
    raw_ms = []
    for _ in range(args.steps):
        raw_ms.append(step_raw())
    red.barrier()
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end)
    # correctness guard: the e2e outputs must be those of the device run
    assert int(si.sum(dtype=np.int64)) == sum_iter, "seam / device mismatch"

    # ---- strong scaling: ONE frame, its tiles dealt to the ranks ----
    strong = None
    if pp_fields:
        mine = [all_tiles[t] for t in multi.tiles_for_rank(n_tiles, rank, world)]
        step_api(mine)                                   # warm-up (staging of this share)
        red.barrier()
        s_ms, s_it = [], 0
        for _ in range(args.steps):
            red.barrier()
            ms, s_it = step_api(mine)
            s_ms.append(ms)
        red.barrier()
        t_strong = red.max(float(np.sum(s_ms))) / args.steps
        it_strong = red.sum(float(s_it))
        strong = {"s_per_frame": t_strong * 1e-3, "ms_per_frame": t_strong,
                  "value": it_strong / (t_strong * 1e-3) / 1e9, "unit": "Gpix-iter/s",
                  "tiles_per_rank": len(mine), "api": "postproc.frame_fields(tiles=rank's share)",
                  "note": "one frame, tiles dealt round-robin to the ranks, max over ranks"}
        assert world > 1 or int(it_strong) == sum_iter

    # ---- the reference's own public call: calc_raw into fresh memmaps ----
    # (core.py:2724-2736 -> compute_rawdata_dev :2515-2554 incl. update_data_mmaps:
    # files created, every tile computed, raw planes written to <dir>/data/*.arr)
    def reset_files():
        data_dir = os.path.join(f.directory, "data")
        for name in (os.listdir(data_dir) if os.path.isdir(data_dir) else []):
            if name.startswith("bench_") and name.endswith(".arr") or name == "bench.report":
                os.unlink(os.path.join(data_dir, name))
        f._calc_data["bench"]["need_new_mmap"] = True
    calc_ms = []
    for k in range(args.steps + 1):                      # first pass: warm-up
        reset_files()
        red.barrier()
        t0 = time.perf_counter()
        f.calc_raw("bench")
        if k > 0:
            calc_ms.append((time.perf_counter() - t0) * 1e3)
    assert int(f.last_stats["sum_stop_iter"]) == sum_iter, "calc_raw / device mismatch"
    reset_files()
    tot_calc = red.max(float(np.sum(calc_ms)))

    tot_dev = red.max(float(np.sum(dev_ms)))
    tot_api = red.max(float(np.sum(api_ms)))
    tot_raw = red.max(float(np.sum(raw_ms)))
    sum_all = red.sum(float(sum_iter))

    ms_per_step = tot_dev / args.steps
    value = sum_all / (ms_per_step * 1e-3) / 1e9
    api_ms_per_step = tot_api / args.steps
    raw_ms_per_step = tot_raw / args.steps

    if rank == 0:
        # ---- roofline of the pixel kernel: FP64 FMA pipe ----
        f_it, f_bla = flops_per_unit(w["kind"], w)
        flops = kstats["n_iter_exec"] * f_it + kstats["n_bla_steps"] * f_bla
        kernel_ms = float(np.mean(dev_ms))
        peak = float(lib.fsb_fp64_peak_tflops(200000))
        achieved = flops / (kernel_ms * 1e-3) / 1e12
        alg_bytes = npts * (16 + n_Z * zdt.itemsize + 4 * n_U + 4 + 1)
        try:
            with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as fh:
                hbm_peak = float(json.load(fh)["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json"
        except Exception:
            hbm_peak, peak_src = 6650., "fallback"
        traffic, traffic_src = None, None
        try:       # dram bytes of one launch from the committed ncu capture (full 4K frames only)
            with open(os.path.join(REPO, "profiles", "traffic.json")) as fh:
                tj = json.load(fh).get(wname)
            if tj and args.nx is None:
                traffic, traffic_src = int(tj["bytes"]), tj["source"]
        except Exception:
            pass
        roofline = {
            "bound": "fp64",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak if peak > 0 else None,
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "measured live: dependent-DFMA chains on all SMs (fsb_fp64_peak_tflops)",
            "flops_per_iter": f_it, "flops_per_bla_step": f_bla,
            "n_iter_exec": int(kstats["n_iter_exec"]),
            "n_bla_steps": int(kstats["n_bla_steps"]),
            "n_rebase": int(kstats["n_rebase"]),
            "n_iter_fast": int(kstats.get("n_iter_fast", 0)),
            "kernel_ms": kernel_ms,
            "hbm": {"achieved": alg_bytes / (kernel_ms * 1e-3) / 1e9,
                    "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                    "algorithmic_bytes": alg_bytes, "peak_source": peak_src},
        }
        cpu_baseline = None
        if with_cpu:
            fc = make_fractal(w, args.nx)
            s = cpu_sample(w, fc, target_s=args.cpu_seconds)
            cpu_baseline = {
                "value": s["iters"] / s["seconds"] / 1e9, "unit": "Gpix-iter/s",
                "cores": host_threads(), "kind": "port",
                "sample": (f"{s['n_tiles']} of {s['n_tiles_total']} tiles, "
                           f"{s['npts']} px, {s['seconds']:.2f} s; oracle "
                           f"C++/OpenMP port (oracle/fs_oracle.cpp)"),
                "s_per_frame_extrapolated": s["seconds"] * s["n_tiles_total"] / s["n_tiles"],
                "tables_s": s["tables_s"],
            }
            cal = numba_calibration(wname)
            if cal:
                cpu_baseline["numba_calibration"] = cal
        h2d_axes = int(axes.axes.nbytes + tile_w.nbytes + tile_h.nbytes)
        d2h_raw = npts * (n_Z * zdt.itemsize + 4 * n_U + 4 + 1)
        if pp_fields:
            n_f32 = len([k for k in pp_out if k not in ("stop_reason", "stop_iter")])
            d2h_api = npts * (4 * n_f32 + 1)
            api_name = ("postproc.frame_fields(fields=%s) -> fsb_frame_run_grid_pp"
                        % (list(pp_fields),))
        else:
            d2h_api = d2h_raw
            api_name = "Fractal.numba_cycle_call(TileAxes, ...) -> fsb_std_run_grid"
        line = {
            "metric": "effective pixel-iterations per second",
            "value": value, "unit": "Gpix-iter/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": bench_config(w, f, world),
            "l2": "flushed between timed steps (192 MiB write); the planes of a "
                  "frame (%d MB) exceed L2" % ((npts * 16 + d2h_raw) // 1000000),
            "build": lib.fsb_build_info().decode(),
            "s_per_frame": ms_per_step * 1e-3,
            "e2e": {"value": sum_all / (api_ms_per_step * 1e-3) / 1e9, "unit": "Gpix-iter/s",
                    "h2d_bytes_per_step": h2d_axes, "d2h_bytes_per_step": d2h_api,
                    "ms_per_step": api_ms_per_step, "s_per_frame": api_ms_per_step * 1e-3,
                    "api": api_name},
            "e2e_raw": {"value": sum_all / (raw_ms_per_step * 1e-3) / 1e9, "unit": "Gpix-iter/s",
                        "h2d_bytes_per_step": h2d_axes, "d2h_bytes_per_step": d2h_raw,
                        "ms_per_step": raw_ms_per_step, "s_per_frame": raw_ms_per_step * 1e-3,
                        "api": "numba_cycle_call(TileAxes, Z, U, stop_reason, stop_iter) "
                               "-> fsb_frame_run_grid, raw planes to pinned host memory"},
            "e2e_api": {"value": sum_all / (tot_calc / args.steps * 1e-3) / 1e9,
                        "unit": "Gpix-iter/s", "ms_per_step": tot_calc / args.steps,
                        "s_per_frame": tot_calc / args.steps * 1e-3,
                        "h2d_bytes_per_step": h2d_axes, "d2h_bytes_per_step": d2h_raw,
                        "api": "Fractal.calc_raw(calc_name): memmaps created, all tiles, raw "
                               "planes written to <directory>/data/*.arr (page cache)"},
            "strong_scaling": strong,
            "gpu_launches": int(args.steps * kstats["n_launches"]),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "setup": {"total_s": setup_s, **{k + "_ms": v for k, v in setup_ms.items()}},
            "sum_stop_iter_per_frame": sum_iter,
        }
    for p in (d_c, d_Z, d_U, d_sr, d_si):
        lib.fsb_dev_free(p)
    for a in (Z, U, sr, si):
        _native.pinned_free(a)
    if perturb:
        frame.close()
    return line


def numba_calibration(wname):
    """ ratio oracle-port / numba-reference measured in the build container on the
    same tile subsets (tools/calibrate_numba.py; the reference is a Python
    package that cannot travel to the GPU box) """
    try:
        with open(os.path.join(REPO, "profiles", "numba_calibration.json")) as fh:
            return json.load(fh).get(wname)
    except Exception:
        return None


def run_movie(args, lib, rank, local_rank, world, dist):
    """ BASELINE config 5: 64 frames 1e-10 -> 1e-2000 at 8K, frames dealt to the
    ranks; one record: seconds per frame and effective Gpix-iter/s of the whole
    job (orbit computed once, outside the timed span, as a movie would cache it) """
    import mpmath
    import fractalshades_b200.models as fsm
    from fractalshades_b200 import movie, settings
    from fractalshades_b200.views import VIEWS
    red = Reducer(dist)
    settings.no_newton = True
    v = VIEWS["deep_julia_2608"]
    dx_end, n_frames, nx = "1e-2000", args.movie_frames, (args.nx or 7680)
    digits = int(-float(mpmath.log10(mpmath.mpf(dx_end)))) + 30
    directory = os.path.join(tempfile.gettempdir(), "fsb_bench_movie")
    seq = movie.ZoomSequence(
        fsm.Perturbation_mandelbrot, directory, x=v["x"][:digits + 20], y=v["y"][:digits + 20],
        dx_start="1e-10", dx_end=dx_end, n_frames=n_frames, nx=nx, xy_ratio=16 / 9.,
        precision=digits,
        calc_kwargs=dict(max_iter=3000000, M_divergence=1e3, epsilon_stationnary=1e-3,
                         BLA_eps=1e-6, interior_detect=False, calc_dzndc=True))
    t_orbit = seq.prepare_orbit(rank)
    red.barrier()
    t0 = time.time()
    recs = seq.render(rank, world, store=False, pp_fields=("cont_iter", "DEM"))
    total_s = time.time() - t0
    red.barrier()
    iters = sum(r["sum_stop_iter"] for r in recs)
    kernel_ms = sum(r["kernel_ms"] for r in recs)
    tmax, isum = red.max(total_s), red.sum(iters)
    if rank != 0:
        return None
    return {"config": {"workload": "deep-zoom movie: %d frames dx 1e-10 -> %s at %dx%d, "
                                   "max_iter 3e6, frames dealt to the ranks" %
                                   (n_frames, dx_end, nx, int(nx * 9 / 16 + 0.5)),
                       "frames_per_step": n_frames},
            "value": isum / tmax / 1e9, "unit": "Gpix-iter/s", "wall_s": tmax,
            "s_per_frame": tmax / n_frames, "orbit_s_once": t_orbit,
            "api": "movie.ZoomSequence.render(pp_fields=('cont_iter', 'DEM'))",
            "rank0": {"frames": len(recs), "kernel_ms_sum": kernel_ms,
                      "setup_s_sum": sum(r["setup_s"] for r in recs),
                      "render_s_sum": sum(r["render_s"] for r in recs),
                      "first": recs[:2], "last": recs[-2:]}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS),
                    help="the workload of the headline line (BASELINE.json's target: config3)")
    ap.add_argument("--movie-frames", type=int, default=64,
                    help="config5 (with --also ...,config5): number of frames of the zoom movie")
    ap.add_argument("--also", default="config1,config2,config4,config5",
                    help="comma list of further workloads reported as sub-records under "
                         "'workloads' (each with its own roofline / e2e / cpu_baseline); "
                         "'none' to skip")
    ap.add_argument("--nx", type=int, default=None, help="debug: override image width")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strict", action="store_true", help="use the -fmad=false build")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    also = [x for x in args.also.split(",") if x and x != "none" and x != args.workload]
    for x in also:
        if x not in WORKLOADS and x != "config5":
            ap.error(f"unknown workload {x}")

    if args.impl == "reference":
        run_reference_arm(args, WORKLOADS[args.workload], rank, world)
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")

    from fractalshades_b200 import _native, settings
    settings.strict_ieee = bool(args.strict)
    os.environ.setdefault("FSB200_DEVICE", str(local_rank))
    lib = _native.cuda_lib()

    with_cpu = (world == 1 and not args.no_cpu_baseline)
    line = run_workload(args, args.workload, lib, rank, local_rank, world, dist, with_cpu)
    subs = {}
    for x in also:
        if x == "config5":
            rec = run_movie(args, lib, rank, local_rank, world, dist)
        else:
            rec = run_workload(args, x, lib, rank, local_rank, world, dist, with_cpu)
        if rec is not None:
            for k in ("n_gpus", "steps", "warmup", "higher_is_better", "vs_baseline",
                      "data", "metric", "unit"):
                rec.pop(k, None)
            subs[x] = rec
    if rank == 0:
        if subs:
            line["workloads"] = subs
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
