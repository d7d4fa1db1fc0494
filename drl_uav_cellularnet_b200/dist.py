"""Multi-GPU plumbing of the step path: one process per GPU, environments sharded by contiguous global-id ranges.

Environments never interact (reference: one env object per worker, main.py:173), so the data path has NO collective:
each rank builds its handle with ``env_offset = lo`` and steps its own envs.  ``torch.distributed`` is used only to
aggregate timings / counters (NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import Sequence

import torch
import torch.distributed as dist

from .env import shard_range  # noqa: F401  (re-exported: the partition rule lives next to the env)


def world() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when launched plainly."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend: str, device: torch.device | None = None) -> None:
    """init_process_group on 127.0.0.1 (the container hostname may not resolve); no-op for world_size 1."""
    _, ws, _ = world()
    if ws <= 1 or dist.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=device)
    else:
        dist.init_process_group(backend)


def _reduce(values: Sequence[float], op, device) -> list[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=op)
    return [float(v) for v in t]


def max_over_ranks(values: Sequence[float], device="cpu") -> list[float]:
    """Element-wise max over ranks (device-timed durations: the job is as slow as its slowest rank)."""
    return _reduce(values, dist.ReduceOp.MAX, device)


def sum_over_ranks(values: Sequence[float], device="cpu") -> list[float]:
    """Element-wise sum over ranks (env counts, step counts)."""
    return _reduce(values, dist.ReduceOp.SUM, device)


def barrier() -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def whole_job_throughput(envs_this_rank: int, steps: int, ms_this_rank: float, device="cpu") -> tuple[float, float]:
    """(env-steps/s of the whole job, max-over-ranks milliseconds): units all ranks processed / slowest rank's time."""
    total_envs = sum_over_ranks([envs_this_rank], device)[0]
    ms = max_over_ranks([ms_this_rank], device)[0]
    return total_envs * steps / (ms * 1e-3), ms
