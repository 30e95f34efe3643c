# -*- coding: utf-8 -*-
"""
Multi-GPU host logic: one process per GPU (torchrun / any launcher), no
data-path collective.  Tiles and movie frames are independent units
(SURVEY.md section 8e): every rank takes its share through the reference's own
`tile_validator` hook (core.py:2521-2523) and writes its tile slabs straight
into the shared memmaps -- that IS the final tile gather.  torch.distributed
(NCCL on the GPU box, gloo in the CPU tests) is only plumbing: a barrier and
the max / sum of the per-rank timing.
"""
import numpy as np


def tiles_for_rank(n_tiles, rank, world, costs=None):
    """ Tile ranks owned by `rank`.  With per-tile cost estimates (e.g. the
    stop_iter sums of a preview frame) tiles are dealt out heaviest first to
    the least loaded rank; otherwise round-robin. """
    if world <= 1:
        return list(range(n_tiles))
    if costs is None:
        return list(range(rank, n_tiles, world))
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    load = np.zeros(world)
    owner = np.zeros(n_tiles, dtype=np.int64)
    for t in order:
        r = int(np.argmin(load))
        owner[t] = r
        load[r] += float(costs[t])
    return [int(t) for t in np.nonzero(owner == rank)[0]]


class RankTiles:
    """ `tile_validator` selecting one rank's tiles (callable on a chunk slice).
    It also carries (rank, world), which lets `Fractal.calc_raw` refuse the one
    unsafe use: several ranks creating the shared memmaps at the same time. """

    def __init__(self, fractal, rank, world, costs=None):
        self.rank, self.world = int(rank), int(world)
        self.fractal = fractal
        self.mine = set(tiles_for_rank(fractal.chunks_count, rank, world, costs))
        self.files_ready = False      # set by calc_raw_sharded after the barrier

    def __call__(self, chunk_slice):
        return self.fractal.chunk_rank(chunk_slice) in self.mine


def tile_validator(fractal, rank, world, costs=None):
    """ A `tile_validator` (core.py:2521-2523) selecting this rank's tiles.

    ORDERING REQUIRED when the ranks share one directory: the report and data
    memmaps must be created by ONE rank before any other rank opens them --
    `Fractal.calc_raw` creates them with mode "w+", which truncates: a late rank
    creating them again would wipe the `done` flags and the slabs the others
    have already written.  Use `calc_raw_sharded`, which implements the
    protocol (rank 0 creates, barrier, the others bind, everybody computes);
    calling `calc_raw` directly with a validator of a rank > 0 while the files
    still have to be created raises. """
    return RankTiles(fractal, rank, world, costs)


def calc_raw_sharded(fractal, calc_name, rank, world, barrier, costs=None):
    """ One frame's tiles over `world` ranks sharing `fractal.directory`:
    rank 0 creates the memmaps, `barrier()` (e.g. torch.distributed.barrier),
    the other ranks bind to the existing files, every rank computes its own
    tiles and writes its slabs (the final tile gather), `barrier()` again.
    Every rank must have run the same calc_std_div(...) before. """
    v = tile_validator(fractal, rank, world, costs)
    if rank == 0:
        fractal.prepare_mmaps(calc_name)
    barrier()
    if rank != 0:
        fractal.bind_mmaps(calc_name)
    v.files_ready = True
    fractal.calc_raw(calc_name, tile_validator=v)
    barrier()
    return v


def frames_for_rank(n_frames, rank, world):
    """ Movie frames (BASELINE config 5) owned by `rank`: interleaved so that
    every rank gets shallow and deep frames alike """
    return list(range(rank, n_frames, max(world, 1)))


def reduce_timing(total_ms, units, all_reduce_max=None, all_reduce_sum=None):
    """ (max over ranks of the timed duration, sum over ranks of the processed
    units): the throughput of the whole job is sum / max.  The two reductions
    are callables supplied by the launcher (bench.py wraps torch.distributed);
    this package itself has no torch dependency. """
    if all_reduce_max is None or all_reduce_sum is None:
        return float(total_ms), float(units)
    return float(all_reduce_max(float(total_ms))), float(all_reduce_sum(float(units)))
