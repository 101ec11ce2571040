"""Multi-GPU plumbing for the inference path: environments shard over ranks (one process per GPU), weights are
replicated, and the only collective is one all-gather of throughput counters after the timed region
(mirrors the reference's rank sharding of evaluation sequences,
mdt/rollout/rollout_long_horizon.py:30-89).  Works with nccl (GPU) and gloo (CPU tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_from_env(backend: str):
    """Initialises torch.distributed from torchrun's environment; no-op for a single process."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(total: int, rank: int, world: int):
    """Contiguous block of environments for this rank (remainder spread over the first ranks)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def barrier():
    if dist.is_available() and dist.is_initialized():
        if dist.get_backend() == "nccl":
            dist.barrier(device_ids=[torch.cuda.current_device()])
        else:
            dist.barrier()


def aggregate_throughput(local_units: float, local_seconds: float, device="cpu"):
    """One all-gather of (units, seconds) per rank.  Whole-job throughput = sum(units) / max(seconds)."""
    t = torch.tensor([float(local_units), float(local_seconds)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        parts = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, t)
        allv = torch.stack(parts).cpu()
    else:
        allv = t.cpu()[None]
    units, secs = float(allv[:, 0].sum()), float(allv[:, 1].max())
    return {"units": units, "seconds": secs, "throughput": units / secs if secs > 0 else 0.0,
            "per_rank": [(float(u), float(s)) for u, s in allv.tolist()]}


def shutdown():
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
