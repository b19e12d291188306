"""Multi-GPU host logic: environments shard trivially.

One process per GPU (torchrun); each rank owns the env slice ``[offset, offset+count)`` of the global batch, a replica
of the car parameters and of the track.  The RNG of random teleports is keyed by the GLOBAL env id, so results do
not depend on the number of ranks.  There is NO per-tick collective; the only exchange of the path is the reduction
of the episode statistics, once per rollout (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os

import numpy as np

STAT_NAMES = ("episodes", "sum_return", "sum_length", "collisions", "offtrack", "stuck", "lowreward", "nan")


def shard_range(total_envs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous env slice of `rank`: (global id of its first env, number of envs).  Remainders go to the low ranks."""
    if world <= 0 or not (0 <= rank < world) or total_envs < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(total_envs, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def env_rank(global_env: int, total_envs: int, world: int) -> int:
    """Inverse of shard_range: which rank owns a global env id."""
    base, rem = divmod(total_envs, world)
    edge = rem * (base + 1)
    return global_env // (base + 1) if global_env < edge else rem + (global_env - edge) // max(base, 1)


def init_from_env(backend: str | None = None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun) and joins the process group when WORLD_SIZE > 1.
    Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            kw = {}
            if backend == "nccl":
                torch.cuda.set_device(local); kw["device_id"] = torch.device("cuda", local)
            dist.init_process_group(backend, **kw)
    return rank, local, world


def reduce_stats(stats, device=None):
    """Sum of the per-rank episode statistics (pd_env_stats' 8 doubles) over all ranks; identity without a group."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(stats, dtype=np.float64))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed durations are reported as the max over ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def summarize(stats) -> dict:
    s = [float(x) for x in stats]
    return {"episodes": s[0], "mean_return": (s[1] / s[0]) if s[0] else None, "mean_length": (s[2] / s[0]) if s[0] else None,
            "collisions": s[3], "offtrack": s[4], "stuck": s[5], "lowreward": s[6], "nan": s[7]}
