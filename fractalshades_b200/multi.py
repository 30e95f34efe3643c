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


def tile_validator(fractal, rank, world, costs=None):
    """ A `tile_validator` for Fractal.calc_raw selecting this rank's tiles """
    mine = set(tiles_for_rank(fractal.chunks_count, rank, world, costs))
    return lambda chunk_slice: fractal.chunk_rank(chunk_slice) in mine


def frames_for_rank(n_frames, rank, world):
    """ Movie frames (BASELINE config 5) owned by `rank`: interleaved so that
    every rank gets shallow and deep frames alike """
    return list(range(rank, n_frames, max(world, 1)))


def reduce_timing(total_ms, units, dist=None, device=None):
    """ (max over ranks of the timed duration, sum over ranks of the processed
    units): the throughput of the whole job is sum / max """
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(total_ms), float(units)
    import torch
    t = torch.tensor([float(total_ms)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])
