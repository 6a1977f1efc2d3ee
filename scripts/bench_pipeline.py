#!/usr/bin/env python
"""BASELINE configs[4]: full pathfinder per-frame pipeline (BGR->gray, grid LK current->previous, vector filter +
danger points, dense Farneback + flow stats, dense-driven filter) on synthetic 3840x2160 batches.
One process per GPU (torchrun), frames sharded with a one-frame halo, per-pair stats gathered to rank 0.

    python scripts/bench_pipeline.py [--height 2160 --width 3840 --pairs 8 --steps 5 --warmup 2]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as tdist
from hackathonopticalflow_b200 import _lib, batch, dist as b2dist, pathfinder, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--pairs", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-dense", action="store_true")
    args = ap.parse_args()
    rank, world, local_rank = b2dist.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    h, w, P = args.height, args.width, args.pairs
    base = synth.sequence(h, w, 5, seed=2000 + rank, gray=False)
    idx = [0, 1, 2, 3, 4, 3, 2, 1] * (P // 8 + 2)
    bgr = torch.from_numpy(np.ascontiguousarray(base[idx[:P + 1]])).to(dev)
    pipe = pathfinder.PathfinderPipeline(h, w, dense=not args.no_dense, chunk_pairs=min(P, 16), device=dev,
                                          side_stream=bool(os.environ.get("B2OF_SIDE")))
    n_frames_global = world * P + 1

    def step():
        out = pipe.run(bgr)
        st = out["stats"] if args.no_dense else torch.cat([out["flow_stats"][:, :4], out["stats"][:, 4:]], 1)
        return out, b2dist.gather_stats(st.contiguous(), n_frames_global, rank, world)

    for _ in range(args.warmup):
        out, stats = step()
    if world > 1:
        tdist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out, stats = step()
    e1.record()
    if world > 1:
        tdist.barrier()
    torch.cuda.synchronize()
    ms = b2dist.max_over_ranks(e0.elapsed_time(e1), dev)
    _lib.profile(True, reset=True)               # per-kernel pass: per-launch events on, dense chunks on one stream
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    prof = _lib.profile()
    _lib.profile(False, reset=True)
    if rank == 0:
        print(json.dumps({
            "metric": "pathfinder_pipeline_frame_pairs_per_sec", "value": world * P * args.steps / (ms * 1e-3),
            "unit": "pairs/s", "n_gpus": world, "config": {"workload": "configs[4] full pipeline", "height": h,
            "width": w, "pairs_per_gpu_per_step": P, "grid_points": int(pipe.points.shape[0]), "dense": not args.no_dense},
            "ms_per_step": ms / args.steps, "n_kept_first_pair": int(out["n_kept"][0]),
            "lk_tracked_fraction": float(out["status"].float().mean()),
            "kernel_ms_per_step": {k: round(v["ms"] / args.steps, 3) for k, v in prof.items()}}))
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
