# -*- coding: utf-8 -*-
"""
ctypes binding of tests/emul/fsb_emul.cu -- the HOST emulation of the lane
state machine of the event-driven pixel kernel (fractalshades_b200/csrc/
fsb_lane.cuh, `lane_step` / `m2_hot_iter`, written `__host__ __device__`).

TEST INFRASTRUCTURE: the product never loads these libraries.  They let the
`-m "not gpu"` suite check the very code the GPU runs, one pixel at a time on
the CPU, against the oracle:

    libfsb_emul_strict.so   -DFSB_STRICT: individually rounded operations =
                            the -fmad=false build; must equal the oracle bit
                            for bit
    libfsb_emul_fma.so      the default build's FMA formulas (libm's exact fma)
"""
import ctypes
import os
import subprocess

import numpy as np

import oracle_lib as ol

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "emul", "fsb_emul.cu")
_CSRC = os.path.join(os.path.dirname(_HERE), "fractalshades_b200", "csrc")
_LIBS = {}

c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_dbl = ctypes.c_double
c_vp = ctypes.c_void_p


class EmulFrame(ctypes.Structure):
    _fields_ = [
        ("L", c_i64), ("Zn_path", c_vp), ("dZndc", c_vp), ("dZndc_e", c_vp),
        ("n_xr", c_i64), ("ref_index_xr", c_vp), ("ref_xr", c_vp), ("ref_xr_e", c_vp),
        ("ref_div_iter", c_i64), ("ref_order", c_i64), ("drift", c_dbl * 2),
        ("drift_e", c_i32), ("lin_scale_e", c_i32), ("lin_scale", c_dbl),
        ("lin_mat", c_dbl * 4), ("M_bla", c_vp), ("r_bla", c_vp), ("bla_len", c_i64),
        ("stages_bla", c_i32), ("xr_detect", c_i32), ("bla_activated", c_i32),
        ("calc_dzndc", c_i32), ("max_iter", c_i64), ("M_divergence_sq", c_dbl),
    ]


def build(strict, force=False):
    """ strict: True / False, or "narrow" = strict with the fp64 lane of the
    Xrange kernels restricted to [2^-40, 2^30] (every guard path gets exercised;
    results must not change: the Xrange form is exact at any magnitude) """
    so = os.path.join(_HERE, "emul", "libfsb_emul_%s.so" % (
        strict if isinstance(strict, str) else ("strict" if strict else "fma")))
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in ("fsb_lane.cuh", "fsb_math.cuh")]
    if (not force and os.path.exists(so)
            and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps)):
        return so
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    # plain g++ (the header is host+device; cuda_runtime.h only supplies the
    # __host__ / __device__ macros and the vector types)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-std=c++17", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++",
           "-I", cuda_inc]
    if strict:
        cmd.append("-DFSB_STRICT")
    else:     # build knobs of the default build under test (e.g. FSB_EMUL_FLAGS=-DFSB_ZZ2=1)
        cmd += os.environ.get("FSB_EMUL_FLAGS", "").split()
    if strict == "narrow":
        cmd += ["-DFSB_FAST_LO=40", "-DFSB_FAST_HI=30"]
    cmd += ["-o", so, _SRC]
    subprocess.check_call(cmd, env=env)
    return so


def lib(strict):
    if strict not in _LIBS:
        _LIBS[strict] = ctypes.CDLL(build(strict))
        assert _LIBS[strict].fsb_emul_strict() == int(bool(strict))
    return _LIBS[strict]


def supported(t):
    """ the variants of the event-driven kernel (fsb200.cu: f->v2) """
    return (t["kind"] == "perturb_M2" and not t.get("nexp") and not t.get("calc_dzndz")
            and not t.get("calc_orbit") and int(t["ref_order"]) >= (1 << 30))


def perturb_m2(t, c_pix, strict=True):
    """ same contract as oracle_lib.perturb_m2 (pixels already projected) """
    keep = []
    f = EmulFrame()
    Zn = ol._c128(t["Zn_path"]); keep.append(Zn)
    f.L = Zn.shape[0]
    f.Zn_path = ol._p(Zn)
    for k, conv in (("dZndc", ol._c128), ("dZndc_e", ol._i32), ("ref_index_xr", ol._i32),
                    ("ref_xr", ol._c128), ("ref_xr_e", ol._i32), ("M_bla", ol._c128),
                    ("r_bla", ol._f64)):
        a = conv(t.get(k)); keep.append(a)
        setattr(f, k, ol._p(a))
    f.n_xr = 0 if t.get("ref_index_xr") is None else len(t["ref_index_xr"])
    f.ref_div_iter = int(t["ref_div_iter"])
    f.ref_order = int(t["ref_order"])
    d = complex(t["drift"])
    f.drift[0], f.drift[1] = d.real, d.imag
    f.drift_e = int(t["drift_e"])
    f.lin_scale = float(t["lin_scale"])
    f.lin_scale_e = int(t["lin_scale_e"])
    lm = np.asarray(t["lin_mat"], np.float64).ravel()
    for i in range(4):
        f.lin_mat[i] = lm[i]
    f.bla_len = int(t.get("bla_len") or 0)
    f.stages_bla = int(t.get("stages_bla") or 0)
    for k in ("xr_detect", "bla_activated", "calc_dzndc"):
        setattr(f, k, int(bool(t.get(k, False))))
    f.max_iter = int(t["max_iter"])
    f.M_divergence_sq = float(t["M_divergence"]) ** 2
    c_pix = ol._c128(c_pix)
    n = c_pix.shape[0]
    Z = np.zeros((ol.nz_m2(t), n), np.complex128)
    U = np.zeros((1, n), np.int32)
    sr = np.full((1, n), -1, np.int8)
    si = np.zeros((1, n), np.int32)
    cnt = np.zeros(8, np.uint64)
    rc = lib(strict).fsb_emul_perturb_m2(
        ctypes.byref(f), c_i64(n), c_vp(ol._p(c_pix)), c_vp(ol._p(Z)), c_vp(ol._p(U)),
        c_vp(ol._p(sr)), c_vp(ol._p(si)), c_vp(ol._p(cnt)))
    if rc != 0:
        raise RuntimeError("fsb_emul_perturb_m2: %d" % rc)
    names = ("n_iter_exec", "n_bla_steps", "n_rebase", "sum_stop_iter", "n_iter_fast",
             "hot_iterations", "event_visits", "failed_guards")
    return Z, U, sr, si, dict(zip(names, (int(x) for x in cnt)))


def perturb(t, c_pix, strict=True, det=True):
    """ projection, loop, dz/dc modifier -- as oracle_lib.perturb """
    pj = t.get("proj")
    pix = ol.project(pj, c_pix, det)
    out = perturb_m2(t, pix, strict)
    ol.apply_modifier(t, out[0], ol.modifier(pj, c_pix, det))
    return out
