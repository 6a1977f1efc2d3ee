"""Frame-sharded data parallelism: one process per GPU, no data-path collective.

A batch of F frames holds F-1 consecutive pairs.  Rank r of R owns the contiguous pairs
``[r*P, min((r+1)*P, F-1))`` with ``P = ceil((F-1)/R)`` and therefore needs frames ``[r*P, (r+1)*P]`` -- a
one-frame halo at the top end that both neighbours load (no exchange).  The only communication is one gather
to rank 0 of a fixed-width per-pair statistics record (8 x float32) after the local work
(the reference itself has no distributed code).
"""
import os

import torch
import torch.distributed as dist

STATS_WIDTH = 8


def pairs_per_rank(n_frames, world):
    n_pairs = max(n_frames - 1, 0)
    return (n_pairs + world - 1) // world if world > 0 else 0


def shard(n_frames, rank, world):
    """(pair_lo, pair_hi, frame_lo, frame_hi): pairs [pair_lo, pair_hi), frames [frame_lo, frame_hi) incl. halo."""
    n_pairs = max(n_frames - 1, 0)
    p = pairs_per_rank(n_frames, world)
    lo = min(rank * p, n_pairs)
    hi = min((rank + 1) * p, n_pairs)
    if hi <= lo:
        return lo, lo, lo, lo
    return lo, hi, lo, hi + 1


def bind_to_local_numa(local_rank):
    """Best effort: pin this process to the CPUs next to its GPU, so pinned host buffers and the copy threads sit
    on the GPU's own NUMA node (matters for the end-to-end path when several ranks stream flow fields to the host)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return sorted(allowed)
    except Exception:
        return None


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            bind_to_local_numa(local_rank)
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


def gather_stats(local_stats, n_frames, rank, world, dst=0):
    """Gather per-pair stats rows to ``dst``.

    local_stats: float32 (n_local_pairs, STATS_WIDTH) on this rank's device, in pair order.
    Returns float32 (F-1, STATS_WIDTH) on ``dst`` (in global pair order), None elsewhere.
    """
    n_pairs = max(n_frames - 1, 0)
    if world == 1:
        return local_stats
    p = pairs_per_rank(n_frames, world)
    padded = torch.zeros((p, STATS_WIDTH), dtype=torch.float32, device=local_stats.device)
    padded[: local_stats.shape[0]] = local_stats
    if rank == dst:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.gather(padded, gather_list=parts, dst=dst)
        return torch.cat(parts, 0)[:n_pairs]
    dist.gather(padded, gather_list=None, dst=dst)
    return None


def max_over_ranks(value, device):
    """Max of a python float over all ranks (used for device-timed regions)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
