# -*- coding: utf-8 -*-
"""
N > 1 host logic on CPU: two processes over the gloo backend.  Each rank takes
its tiles through the tile_validator hook, writes its slabs into the SHARED
memmaps with the product's own writer (the "final tile gather"), and the
timing reduction is max / sum over ranks.  No GPU: the tile payload is a
deterministic function of the pixel coordinates standing in for kernel output.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

from fractalshades_b200 import multi

HERE = os.path.dirname(os.path.abspath(__file__))

WORKER = r'''
import os, sys
sys.path.insert(0, os.path.dirname(%(here)r)); sys.path.insert(0, %(here)r)
import numpy as np
import torch.distributed as dist
import fractalshades_b200 as fsb, fractalshades_b200.models as fsm
from fractalshades_b200 import multi
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
workdir = sys.argv[1]
def setup():
    f = fsm.Mandelbrot(workdir)
    f.zoom(x=-1., y=0., dx=5., nx=500, xy_ratio=1.25, theta_deg=0.)
    # bind the calculation without creating a device frame
    kw = dict(calc_name="c", subset=None, max_iter=100, M_divergence=1000., epsilon_stationnary=1e-3)
    f.calc_std_div(**kw)
    return f
# the protocol of multi.calc_raw_sharded: rank 0 creates the directory, the
# parameter files and the memmaps, barrier, the others bind to them
if rank == 0:
    f = setup()
    f.prepare_mmaps("c")
    dist.barrier()
else:
    dist.barrier()
    f = setup()
    f.bind_mmaps("c")
dist.barrier()
validator = multi.tile_validator(f, rank, world)
assert (validator.rank, validator.world) == (rank, world)
n_mine = 0
for cs in f.chunk_slices():
    if not validator(cs):
        continue
    (c_pix, Z, U, sr, si), _ = f.get_cycling_dep_args("c", cs)
    Z[0] = c_pix; Z[1] = 2 * c_pix; Z[2] = rank
    si[0] = (np.abs(c_pix) * 1000).astype(np.int32); sr[0] = 1
    f.update_data_mmaps("c", cs, Z, U, sr, si)
    f.update_report_mmap("c", cs)
    n_mine += 1
import torch
def _red(op):
    def f_(v):
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t[0])
    return f_
tmax, usum = multi.reduce_timing(10.0 * (rank + 1), n_mine, _red(dist.ReduceOp.MAX), _red(dist.ReduceOp.SUM))
dist.barrier()
if rank == 0:
    rep = np.array(f.get_report_memmap("c", mode="r"))
    assert rep[:, 3].all(), rep[:, 3]
    Zm = f.get_data_memmap("c", "Z", mode="r"); sim = f.get_data_memmap("c", "stop_iter", mode="r")
    for r_, cs in enumerate(f.chunk_slices()):
        c_pix = np.ravel(f.chunk_pixel_pos(cs, False, None))
        beg, end = rep[r_, 0], rep[r_, 1]
        assert np.array_equal(Zm[0, beg:end], c_pix)
        assert np.all(Zm[2, beg:end].real == (r_ %% world))
        assert np.array_equal(sim[0, beg:end], (np.abs(c_pix) * 1000).astype(np.int32))
    assert tmax == 10.0 * world and usum == f.chunks_count, (tmax, usum)
    print("MULTI_OK", f.chunks_count, tmax, usum)
dist.destroy_process_group()
'''


def test_tile_partition_properties():
    for n, w in ((220, 8), (858, 4), (16, 2), (5, 8), (1, 2)):
        seen = []
        for r in range(w):
            seen += multi.tiles_for_rank(n, r, w)
        assert sorted(seen) == list(range(n))
    costs = np.arange(220)[::-1] ** 2
    loads = [sum(costs[t] for t in multi.tiles_for_rank(220, r, 8, costs)) for r in range(8)]
    assert max(loads) / (sum(loads) / 8) < 1.02          # balanced within 2 %
    assert multi.frames_for_rank(64, 3, 8) == list(range(3, 64, 8))
    assert multi.reduce_timing(5., 7.) == (5., 7.)


def test_calc_raw_refuses_concurrent_creation():
    """ a rank > 0 validator must not be the one that creates (truncates) the
    shared memmaps: calc_raw raises before touching any file or the GPU """
    import pytest
    import fractalshades_b200.models as fsm
    f = fsm.Mandelbrot(tempfile.mkdtemp())
    f.zoom(x=-1., y=0., dx=5., nx=64, xy_ratio=1., theta_deg=0.)
    f.calc_std_div(calc_name="c", subset=None, max_iter=10, M_divergence=100.,
                   epsilon_stationnary=1e-3)
    v = multi.tile_validator(f, 1, 2)
    assert f._calc_data["c"]["need_new_mmap"]
    with pytest.raises(RuntimeError, match="calc_raw_sharded"):
        f.calc_raw("c", tile_validator=v)
    assert not os.path.exists(f.report_path("c"))
    # binding to files that do not exist fails loudly too
    with pytest.raises(FileNotFoundError):
        f.bind_mmaps("c")
    f.prepare_mmaps("c")
    f.bind_mmaps("c")
    assert not f._calc_data["c"]["need_new_mmap"]


def test_two_ranks_gloo_shared_memmaps():
    d = tempfile.mkdtemp()
    script = os.path.join(d, "worker.py")
    open(script, "w").write(WORKER % {"here": HERE})
    import socket
    with socket.socket() as sk:          # a free rendezvous port
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
               WORLD_SIZE="2", OMP_NUM_THREADS="1")
    procs = []
    for rank in range(2):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, script, os.path.join(d, "work")],
                                      env=e, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "MULTI_OK" in outs[0]
