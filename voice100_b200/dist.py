"""Multi-GPU plumbing for the utterance-sharded path.

Utterances are independent, so ranks share NOTHING on the data path: each rank runs the whole pipeline on its own
shard with a full weight replica.  torch.distributed is used only to (a) agree on the shards and (b) reduce timings
and counters after the timed region (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Assign utterance indices to ranks so that the total number of samples per rank is balanced:
    sort by length (longest first), then deal greedily to the currently lightest rank.  Deterministic,
    every rank computes the same answer locally (no communication)."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    shards: List[List[int]] = [[] for _ in range(world_size)]
    load = [0] * world_size
    for i in order:
        r = min(range(world_size), key=lambda j: (load[j], j))
        shards[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(s) for s in shards]


def max_over_ranks(value: float, device="cpu") -> float:
    """Step time of the job = the slowest rank's device time."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    """Whole-job unit count (audio-seconds, launches, ...)."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def job_throughput(units_this_rank: float, seconds_this_rank: float, device="cpu") -> float:
    """Aggregate throughput: all ranks' units / the slowest rank's time."""
    return sum_over_ranks(units_this_rank, device) / max_over_ranks(seconds_this_rank, device)


def bind_to_gpu_numa(device_index: int) -> bool:
    """Pin this process to the CPU cores NVML reports as local to GPU `device_index`, so that pinned host buffers
    allocated afterwards are first-touched on that NUMA node and H2D copies of the ranks do not all cross one
    socket's memory controllers.  Best effort: returns False when NVML or the affinity call is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False
