# -*- coding: utf-8 -*-
"""
In-tree build of the native libraries (explicit nvcc / gcc invocations; the
built .so files travel with the repository snapshot to the GPU box).

    libfsb200.so         nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
    libfsb200_strict.so  same + -fmad=false -DFSB_STRICT (IEEE-strict variant)
    libfsb200_orbit.so   gcc, links the MPFR / MPC runtime libraries
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")

NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-O3", "-Xcompiler", "-fPIC", "-shared",
              "-cudart", "static"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return res.stdout


def cuda_sources():
    return [os.path.join(CSRC, f) for f in
            ("fsb200.cu", "fsb_kernels.cuh", "fsb_lane.cuh", "fsb_math.cuh")] + [
            os.path.join(INCLUDE, "fsb200.h")]


def build_cuda(strict=False, force=False, verbose=False):
    out = os.path.join(PKG, "libfsb200_strict.so" if strict else "libfsb200.so")
    if not force and not _stale(out, cuda_sources()):
        return out
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS
    if strict:
        cmd += ["-fmad=false", "-DFSB_STRICT"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    cmd += ["-o", out, os.path.join(CSRC, "fsb200.cu")]
    log = _run(cmd)
    if verbose:
        print(log)
    return out


def build_orbit(force=False):
    out = os.path.join(PKG, "libfsb200_orbit.so")
    src = [os.path.join(CSRC, "fp_orbit.c"), os.path.join(INCLUDE, "fsb200_orbit.h")]
    if not force and not _stale(out, src):
        return out
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    _run([gcc, "-O2", "-fPIC", "-shared", "-o", out, src[0],
          "-l:libmpc.so.3", "-l:libmpfr.so.6", "-l:libgmp.so.10", "-lm"])
    return out


def build_all(force=False):
    return [build_orbit(force), build_cuda(False, force), build_cuda(True, force)]


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv):
        print(p)
