# -*- coding: utf-8 -*-
"""
BASELINE config 5: deep-zoom movie, frames sharded over the ranks.
    python tools/movie_bench.py [--frames 64] [--nx 7680] [--dx-end 1e-2000]
    torchrun --nproc-per-node N tools/movie_bench.py ...
Prints one JSON line (rank 0): seconds per frame, effective Gpix-iter/s.
"""
import argparse, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--nx", type=int, default=7680)
    ap.add_argument("--dx-start", default="1e-10")
    ap.add_argument("--dx-end", default="1e-2000")
    ap.add_argument("--max-iter", type=int, default=3000000)
    ap.add_argument("--dir", default=None)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch, torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
    os.environ.setdefault("FSB200_DEVICE", str(local_rank))
    import mpmath
    import fractalshades_b200 as fsb
    import fractalshades_b200.models as fsm
    from fractalshades_b200 import movie, multi, settings
    settings.no_newton = True      # reference point = image centre
    v = fsb.VIEWS["deep_julia_2608"]
    digits = int(-float(mpmath.log10(mpmath.mpf(args.dx_end)))) + 30
    directory = args.dir or os.path.join(tempfile.gettempdir(), "fsb_movie")
    seq = movie.ZoomSequence(
        fsm.Perturbation_mandelbrot, directory, x=v["x"][:digits + 20], y=v["y"][:digits + 20],
        dx_start=args.dx_start, dx_end=args.dx_end, n_frames=args.frames, nx=args.nx,
        xy_ratio=16 / 9., precision=digits,
        calc_kwargs=dict(max_iter=args.max_iter, M_divergence=1e3, epsilon_stationnary=1e-3,
                         BLA_eps=1e-6, interior_detect=False, calc_dzndc=True))
    t_orbit = seq.prepare_orbit(rank)
    if dist is not None:
        dist.barrier()
    t0 = time.time()
    recs = seq.render(rank, world, store=False)
    total_s = time.time() - t0
    iters = sum(r["sum_stop_iter"] for r in recs)
    kernel_ms = sum(r["kernel_ms"] for r in recs)
    red_max = red_sum = None
    if dist is not None:
        import torch

        def _red(op):
            def f_(v):
                t = torch.tensor([v], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=op)
                return float(t[0])
            return f_
        red_max, red_sum = _red(dist.ReduceOp.MAX), _red(dist.ReduceOp.SUM)
    tmax, isum = multi.reduce_timing(total_s * 1e3, iters, red_max, red_sum)
    if rank == 0:
        print(json.dumps({
            "workload": f"deep-zoom movie {args.frames} frames {args.dx_start}->{args.dx_end} at {args.nx}px",
            "n_gpus": world, "orbit_s": t_orbit, "wall_s": tmax * 1e-3,
            "s_per_frame": tmax * 1e-3 / args.frames, "value": isum / (tmax * 1e-3) / 1e9,
            "unit": "Gpix-iter/s", "rank0_frames": len(recs),
            "rank0_kernel_ms_sum": kernel_ms,
            "rank0_setup_s_sum": sum(r["setup_s"] for r in recs),
            "rank0_render_s_sum": sum(r["render_s"] for r in recs),
            "frames": recs[:4] + recs[-2:]}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
