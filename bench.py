#!/usr/bin/env python
"""Headline benchmark: Farneback frame-pairs/sec @1920x1080 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the dense-flow hot path over one batch of synthetic 1080p frames per GPU: PAIRS consecutive
pairs (PAIRS+1 device-resident uint8 gray frames, the one-frame halo included) -> PAIRS float32 flow fields in HBM,
plus the per-pair flow statistics and (N>1) their gather to rank 0.  Weak scaling: per-GPU work is fixed.

value  = pairs all ranks processed / max-over-ranks device time (CUDA events), inputs resident in HBM.
e2e    = the same metric through the cv2-compatible host call (pinned host frames in, host flow out; the H2D and
         D2H copies are inside the timed region).
roofline = the dominant kernel (finest-level fused iteration) timed with CUDA events around every launch on its
         stream, over a second pass of the same K steps inside this run (`kernel_timing_pass`): the timed region walks
         a chunk as several ranges of pairs on several streams, where a launch's events would time interleaved
         kernels; with the events on the library keeps the chunk on one stream.
cpu_baseline = the reference's own CPU path (cv2.calcOpticalFlowFarneback, one cv2 thread per pair over all
         host cores) on a bounded sample of the same workload.  Reported beside, not the target.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 1080, 1920
PAIRS_PER_GPU = 64          # pairs per step per GPU (65 frames = 135 MB of input, > L2)
CHUNK_PAIRS = 64            # pairs in flight per pass inside the engine (13 GB of scratch)
E2E_PAIRS = 64              # pairs per end-to-end step (135 MB H2D, 1062 MB D2H)
PARAMS = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)  # DenseOF.py:127-128
WORKLOAD = "configs[2]: DenseOF Farneback on synthetic 1920x1080 frame-pair batches, reference parameters"


def level_pixels(w, h, levels=3, scale=0.5):
    tot, n0 = 0, w * h
    for k in range(levels + 1):
        tot += round(w * scale ** k) * round(h * scale ** k)
    return n0, tot


def stream_bytes_per_pair():
    """B_stream = 196*S - 7*N for consecutive frames (per-frame work reused)."""
    n, s = level_pixels(W, H)
    return 196 * s - 7 * n


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if not (t0 <= ts <= t1 + 0.2):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def synthetic_frames(n_frames, seed):
    """uint8 (n_frames,H,W) gray: a 17-frame seeded flight played forwards and backwards."""
    import numpy as np
    from hackathonopticalflow_b200 import synth
    base = synth.sequence(H, W, 17, seed=seed)
    idx, i, d = [], 0, 1
    while len(idx) < n_frames:
        idx.append(i)
        if i + d < 0 or i + d > 16:
            d = -d
        i += d
    return np.ascontiguousarray(base[idx])


def cpu_reference_rate(n_pairs, workers, seed=999):
    """cv2.calcOpticalFlowFarneback, one cv2 thread per pair over `workers` host threads -> pairs/s."""
    from oracle import cv2_reference as ref
    frames = synthetic_frames(n_pairs + 1, seed)
    ref.farneback_pool(frames[:3], workers)  # touch code paths / page in
    t0 = time.perf_counter()
    ref.farneback_pool(frames, workers)
    dt = time.perf_counter() - t0
    return n_pairs / dt, dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (cv2) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import cv2_reference as ref
    cores = ref.host_cores()
    per_step = max(cores, 2)
    frames = synthetic_frames(per_step + 1, 999)
    for _ in range(args.warmup):
        ref.farneback_pool(frames[: min(len(frames), cores + 1)], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.farneback_pool(frames, cores)
    dt = time.perf_counter() - t0
    v = per_step * args.steps / dt
    line = {"impl": "reference", "metric": "farneback_frame_pairs_per_sec_1080p", "value": v, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "height": H, "width": W, "pairs_per_step": per_step,
                       "path": "cv2.calcOpticalFlowFarneback (opencv-python wheel), ThreadPoolExecutor over pairs, "
                               "cv2.setNumThreads(1)"},
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "reference",
                             "sample": f"{per_step} pairs per step x {args.steps} steps, 1080p synthetic"},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def decode_png(buf):
    """PNG bytes (numpy uint8) -> uint8 (H,W).  Fixture decoding only."""
    from io import BytesIO
    import numpy as np
    from PIL import Image
    return np.array(Image.open(BytesIO(buf.tobytes())))


def pcie_ceiling(dev, h2d_bytes, d2h_bytes, world, reps=3):
    """Bare pinned copies of one end-to-end step's bytes (H2D of the frames, D2H of the flow fields, on two streams,
    every rank at once): the rate no implementation that returns float32 flow fields to the host can exceed."""
    import torch
    import torch.distributed as tdist
    from hackathonopticalflow_b200 import dist as b2dist
    hin = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    din = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    dout = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    best = None
    for i in range(reps + 1):
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        dt = b2dist.max_over_ranks(time.perf_counter() - t0, dev)
        if i > 0:
            best = dt if best is None else min(best, dt)
    return best


def real_1080p_record(batch, dev, pairs):
    """Real footage at the clips' native 1080p (BASELINE configs[0]): the four committed full-resolution pairs, one per
    reference clip (tests/golden/real_1080p.npz + real_sweep.npz), each played a b a b ... for a quarter of the step's
    pairs (flows a->b and b->a alternate; the pair where one clip hands over to the next is a scene cut and is
    computed like any other): throughput on real content (large motions, discontinuities, dark footage) and the error
    of each clip's first pair against the committed cv2 result."""
    import numpy as np
    import torch
    gold = os.path.join(ROOT, "tests", "golden")
    if not os.path.exists(os.path.join(gold, "real_1080p.npz")):
        return {"unavailable": "tests/golden/real_1080p.npz missing"}
    z = np.load(os.path.join(gold, "real_1080p.npz"))
    clips = [(decode_png(z["png0"]), decode_png(z["png1"]), z["flow_s8"], "real_1080p.npz", None)]
    if os.path.exists(os.path.join(gold, "real_sweep.npz")):
        zs = np.load(os.path.join(gold, "real_sweep.npz"))
        clips[0] = clips[0][:4] + (zs["stable_3"],)
        clips += [(decode_png(zs[f"png0_{i}"]), decode_png(zs[f"png1_{i}"]), zs[f"flow_s8_{i}"], f"real_sweep.npz[{i}]",
                   zs[f"stable_{i}"]) for i in range(3)]
    per = max(2, (pairs // len(clips)) & ~1)          # frames per clip (even: every clip starts on its a frame)
    seq, first = [], []
    for a, b, _, _, _ in clips:
        first.append(len(seq))
        seq += [a, b] * (per // 2)
    seq = (seq + [clips[-1][0], clips[-1][1]] * (pairs // 2 + 1))[:pairs + 1]
    frames = torch.from_numpy(np.ascontiguousarray(np.stack(seq))).to(dev)
    eng = batch.FarnebackEngine(H, W, chunk_pairs=pairs, device=dev, **PARAMS)
    flow = torch.empty((pairs, H, W, 2), dtype=torch.float32, device=dev)
    for _ in range(2):
        eng.flow_sequence(frames, flow)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        eng.flow_sequence(frames, flow)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    per_clip = []
    for (a, b, want, src, bits), k in zip(clips, first):
        if k >= pairs:
            break
        got = flow[k, ::8, ::8].cpu().numpy().astype(np.float64)
        d = np.sqrt(((got - want.astype(np.float64)) ** 2).sum(-1))
        mag = np.sqrt((want.astype(np.float64) ** 2).sum(-1))
        rec = {"source": src, "epe_mean_vs_cv2": float(d.mean()), "epe_max_vs_cv2": float(d.max()),
               "flow_mean_px": float(mag.mean()), "flow_max_px": float(mag.max())}
        if bits is not None:
            stable = np.unpackbits(bits)[:d.size].reshape(d.shape).astype(bool)
            rec.update({"cv2_stable_fraction": float(stable.mean()), "epe_max_on_cv2_stable_px": float(d[stable].max()),
                        "fraction_over_0.5px": float((d > 0.5).mean())})
        per_clip.append(rec)
    # the same sequence with the library's blocked horizontal sums (B2OF_FARNEBACK_BLOCKED_SUMS, include/b2of.h)
    eng_b = batch.FarnebackEngine(H, W, chunk_pairs=pairs, device=dev, **{**PARAMS, "flags": PARAMS.get("flags", 0) | 0x10000})
    for _ in range(2):
        eng_b.flow_sequence(frames, flow)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        eng_b.flow_sequence(frames, flow)
    e1.record()
    torch.cuda.synchronize()
    ms_b = e0.elapsed_time(e1) / n
    blocked = []
    for (a, b, want, src, bits), k in zip(clips, first):
        if k >= pairs or bits is None:
            continue
        got = flow[k, ::8, ::8].cpu().numpy().astype(np.float64)
        d = np.sqrt(((got - want.astype(np.float64)) ** 2).sum(-1))
        stable = np.unpackbits(bits)[:d.size].reshape(d.shape).astype(bool)
        blocked.append({"source": src, "epe_mean_vs_cv2": float(d.mean()), "epe_max_on_cv2_stable_px": float(d[stable].max()),
                        "fraction_over_0.5px": float((d > 0.5).mean())})
    del eng, eng_b, flow, frames
    return {"pairs_per_s": pairs / (ms * 1e-3), "ms_per_step": ms, "pairs_per_step": pairs,
            "blocked_sums": {"pairs_per_s": pairs / (ms_b * 1e-3), "clips": blocked,
                             "note": "flags | B2OF_FARNEBACK_BLOCKED_SUMS: horizontal box sums per block of 15 instead "
                                     "of sliding (each output sums only its own inputs)"},
            "epe_mean_vs_cv2": max(c["epe_mean_vs_cv2"] for c in per_clip),
            "epe_max_vs_cv2": max(c["epe_max_vs_cv2"] for c in per_clip),
            "flow_mean_px": max(c["flow_mean_px"] for c in per_clip),
            "flow_max_px": max(c["flow_max_px"] for c in per_clip),
            "epe_max_on_cv2_stable_px": max(c.get("epe_max_on_cv2_stable_px", c["epe_max_vs_cv2"]) for c in per_clip),
            "clips": per_clip,
            "conditioning": "cv2_stable = pixels where cv2's own flow moves < 0.05 px under one grey level of noise on "
                            "0.1 % of the input and between its SIMD / plain builds (tests/golden/make_golden_sweep.py); "
                            "elsewhere cv2 differs from itself by up to 2.4 px on these pairs",
            "source": "one full-resolution pair of each of the reference's four clips (tests/golden/real_1080p.npz, "
                      "real_sweep.npz), cv2 flow sampled every 8th px; top-level error figures are the worst clip's"}


def pipeline_4k_record(rank, world, dev, pairs=8, steps=3):
    """BASELINE configs[4]: the full pathfinder per-frame pipeline on synthetic 3840x2160 batches, every rank its own
    shard, per-pair stats gathered to rank 0 (outside the timed headline)."""
    import numpy as np
    import torch
    import torch.distributed as tdist
    from hackathonopticalflow_b200 import _lib, dist as b2dist, pathfinder, synth
    h, w = 2160, 3840
    base = synth.sequence(h, w, 3, seed=2000 + rank, gray=False)
    idx = ([0, 1, 2, 1] * (pairs // 4 + 1))[:pairs + 1]
    bgr = torch.from_numpy(np.ascontiguousarray(base[idx])).to(dev)
    pipe = pathfinder.PathfinderPipeline(h, w, dense=True, chunk_pairs=pairs, device=dev)
    n_frames_global = world * pairs + 1

    def step():
        out = pipe.run(bgr)
        st = torch.cat([out["flow_stats"][:, :4], out["stats"][:, 4:]], 1).contiguous()
        return out, b2dist.gather_stats(st, n_frames_global, rank, world)

    for _ in range(2):
        out, st = step()
    if world > 1:
        tdist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out, st = step()
    e1.record()
    if world > 1:
        tdist.barrier()
    torch.cuda.synchronize()
    ms = b2dist.max_over_ranks(e0.elapsed_time(e1), dev) / steps
    _lib.profile(True, reset=True)               # per-kernel pass (one stream, per-launch events), as in main()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    prof = _lib.profile()
    _lib.profile(False, reset=True)
    rec = {"pairs_per_s": world * pairs / (ms * 1e-3), "ms_per_step": ms, "pairs_per_gpu_per_step": pairs,
           "n_gpus": world, "height": h, "width": w, "grid_points": int(pipe.points.shape[0]),
           "n_kept_first_pair": int(out["n_kept"][0]), "lk_tracked_fraction": float(out["status"].float().mean()),
           "stats_rows_on_rank0": int(st.shape[0]) if st is not None else None,
           "kernel_ms_per_step": {k: round(v["ms"] / steps, 3) for k, v in prof.items()},
           "workload": "configs[4]: BGR -> gray -> 9216-point grid LK 45x45 (current -> previous) -> vector filter + "
                       "danger points + dense Farneback + stats, synthetic 3840x2160"}
    del pipe, bgr, out
    torch.cuda.empty_cache()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU)
    ap.add_argument("--chunk", type=int, default=CHUNK_PAIRS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sub-records measured outside the timed headline (real_1080p, pipeline_4k, xrank_check)")
    ap.add_argument("--separate-stats", action="store_true",
                    help="A/B only: per-pair statistics by a second pass over the flow fields (b2of_flow_stats_dev)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as tdist
    from hackathonopticalflow_b200 import _lib, batch, cv2compat, dist as b2dist

    rank, world, local_rank = b2dist.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    P = args.pairs
    n_frames_global = world * P + 1
    lo, hi, flo, fhi = b2dist.shard(n_frames_global, rank, world)   # P pairs, P+1 frames (one-frame halo)
    frames_host = synthetic_frames(fhi - flo, 1000 + rank)
    frames = torch.from_numpy(frames_host).to(dev)
    eng = batch.FarnebackEngine(H, W, chunk_pairs=args.chunk, device=dev, **PARAMS)
    flow = torch.empty((P, H, W, 2), dtype=torch.float32, device=dev)
    pair_stats = torch.empty((P, 8), dtype=torch.float32, device=dev)
    lib = _lib.lib()

    def step():
        # flow fields + per-pair statistics (reduced inside the last iteration kernel), then the gather to rank 0
        if args.separate_stats:
            eng.flow_sequence(frames, flow)
            return b2dist.gather_stats(batch.flow_stats(flow), n_frames_global, rank, world)
        eng.flow_sequence(frames, flow, stats=pair_stats)
        return b2dist.gather_stats(pair_stats, n_frames_global, rank, world)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        stats = step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = lib.b2of_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        stats = step()
    e1.record()
    barrier()
    t1 = time.perf_counter()
    ms = b2dist.max_over_ranks(e0.elapsed_time(e1), dev)
    launches = lib.b2of_launch_count() - launches0
    clocks = sampler.stop(t0, t1) if sampler else None
    # Per-kernel pass: the same K steps again with the library's per-launch CUDA events on.  In the timed region
    # above the library walks a 64-pair chunk as eight ranges of pairs on eight streams (a partial last wave of one
    # launch is filled by the other ranges' CTAs), where a launch's events would time several ranges' kernels
    # interleaved; with the events on, the chunk runs as one range on one stream and every launch is timed alone.
    _lib.profile(True, reset=True)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        step()
    p1.record()
    barrier()
    serial_ms = p0.elapsed_time(p1) / args.steps
    prof = _lib.profile()
    _lib.profile(False, reset=True)
    if world > 1:
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        tdist.all_reduce(lt)
        launches = int(lt.item())
    value = world * P * args.steps / (ms * 1e-3)

    # ---- end to end through the cv2-compatible host call: pinned host frames in, host flow out ----
    E = min(E2E_PAIRS, P)
    seq_h = torch.from_numpy(frames_host[:E + 1]).pin_memory()
    flow_h = torch.empty((E, H, W, 2), dtype=torch.float32).pin_memory()
    e2e_steps = 0 if args.no_e2e else max(3, min(args.steps, 10))
    for _ in range(0 if args.no_e2e else 2):
        cv2compat.calcOpticalFlowFarnebackSequence(seq_h.numpy(), flow=flow_h.numpy(), **PARAMS)
    barrier()
    te0 = time.perf_counter()
    for _ in range(e2e_steps):
        cv2compat.calcOpticalFlowFarnebackSequence(seq_h.numpy(), flow=flow_h.numpy(), **PARAMS)
    torch.cuda.synchronize()
    e2e_ms = b2dist.max_over_ranks((time.perf_counter() - te0) * 1e3, dev)
    e2e_value = world * E * e2e_steps / (e2e_ms * 1e-3) if e2e_steps else None
    ceiling_s = None if args.no_e2e else pcie_ceiling(dev, (E + 1) * H * W, E * H * W * 8, world)

    # ---- outside the timed headline -------------------------------------------------------------------------------
    xrank = None
    if world > 1 and not args.no_extras:
        # rank 0 regenerates the LAST rank's shard (seeds are 1000 + rank), runs the same step on its own GPU and
        # compares its statistics rows bit for bit with the rows that arrived through the NCCL gather
        if rank == 0:
            other = torch.from_numpy(synthetic_frames(P + 1, 1000 + world - 1)).to(dev)
            mine = torch.empty((P, 8), dtype=torch.float32, device=dev)
            eng.flow_sequence(other, flow, stats=mine)
            got = stats[(world - 1) * P:world * P]
            xrank = "ok" if torch.equal(mine, got) else "MISMATCH: %d of %d rows differ" % (
                int((mine != got).any(1).sum()), P)
            del other
    extras = {}
    if not args.no_extras:
        del flow
        torch.cuda.empty_cache()
        if rank == 0:
            extras["real_1080p"] = real_1080p_record(batch, dev, P)
        extras["pipeline_4k"] = pipeline_4k_record(rank, world, dev)

    if rank != 0:
        if world > 1:
            tdist.barrier()
            tdist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak_gbs()
    dom = prof.get("fb_iter_finest", {"ms": 0.0, "launches": 0, "bytes": 0.0})
    achieved = dom["bytes"] / (dom["ms"] * 1e-3) / 1e9 if dom["ms"] > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            per_pair = json.load(f).get("fb_iter_finest_dram_bytes_per_pair_launch")
            traffic = per_pair * min(args.chunk, P) if per_pair else None
    except Exception:
        pass
    b_stream = stream_bytes_per_pair()
    path_gbs = value / world * b_stream / 1e9
    total_ms = sum(v["ms"] for v in prof.values())
    line = {
        "metric": "farneback_frame_pairs_per_sec_1080p", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "height": H, "width": W, "pairs_per_gpu_per_step": P,
                   "frames_per_gpu": P + 1, "chunk_pairs": args.chunk, "input_form": "consecutive frames "
                   "(per-frame work reused, B_stream accounting)", "sharding": f"frames x{world} contiguous + 1-frame halo",
                   "l2": "inputs (135 MB/GPU) and intermediates (13 GB/GPU) exceed the 126 MB L2; no explicit flush",
                   "step_includes": "flow_sequence with the per-pair statistics reduced inside its last kernel + gather "
                                    "of the stats rows to rank 0"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": (E + 1) * H * W,
                "d2h_bytes_per_step": E * H * W * 8, "pairs_per_step": E, "steps": e2e_steps,
                "api": "cv2compat.calcOpticalFlowFarnebackSequence -> b2of_farneback_sequence_host (pinned host buffers; "
                       "H2D of the frames and D2H of every flow field inside the timed region)",
                "pcie_ceiling_pairs_per_s": world * E / ceiling_s if ceiling_s else None,
                "pcie_ceiling_note": "bare pinned H2D + D2H of one step's bytes on two streams, every rank at once "
                                     "(max over ranks): the ranks of a node share the host's PCIe uplinks, so this "
                                     "ceiling, not NCCL, is what the host-buffer number scales with"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "fb_iter_ws @ finest level (warp-specialised fused UpdateMatrices + 15x15 box + 2x2 solve)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic,
                     "traffic_source": "profiles/roofline_traffic.json: ncu dram__bytes_read + write of the shipped kernel "
                                       "(one --set full capture per kernel change), scaled to this launch's pairs; a "
                                       "constant, not re-measured in this run",
                     "peak_source": peak_src, "launches": dom["launches"],
                     "avg_launch_ms": dom["ms"] / dom["launches"] if dom["launches"] else None,
                     "algorithmic_bytes_per_launch": dom["bytes"] / dom["launches"] if dom["launches"] else None,
                     "share_of_step": dom["ms"] / total_ms if total_ms else None,
                     "timing": "CUDA events around every launch on its stream, over a second pass of the same K steps "
                               "with the chunk on one stream (kernel_timing_pass)"},
        "path_roofline": {"bytes_per_pair": b_stream, "achieved": path_gbs, "unit": "GB/s",
                          "frac": path_gbs / peak, "note": "whole dense-flow path per GPU, SURVEY 8(d) B_stream"},
        "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
        "kernel_timing_pass": {"ms_per_step": serial_ms, "sum_of_kernels_ms": total_ms / args.steps,
                               "note": "one stream, per-launch events on; the timed region (ms_per_step) walks the "
                                       "chunk as up to eight ranges of pairs on as many streams with the events off"},
    }
    if world == 1 and not args.no_cpu:
        from oracle import cv2_reference as ref
        cores = ref.host_cores()
        n = min(max(2 * cores, 4), 64)
        v, dt = cpu_reference_rate(n, cores)
        line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "reference",
                                "sample": f"{n} consecutive 1080p synthetic pairs, cv2.calcOpticalFlowFarneback, "
                                          f"one cv2 thread per pair x {cores} threads, {dt:.1f} s wall"}
    else:
        line["cpu_baseline"] = None
    if stats is not None:
        line["stats_rows_on_rank0"] = int(stats.shape[0])
    if xrank is not None:
        line["xrank_check"] = xrank
    line.update(extras)
    print(json.dumps(line))
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
