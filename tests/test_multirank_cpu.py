"""N>1 host logic under gloo, world_size 2, on the CPU: shard agreement, balance, and the max-over-ranks /
sum-over-ranks aggregation bench.py reports with."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from voice100_b200 import synth
from voice100_b200.dist import job_throughput, max_over_ranks, shard_utterances, sum_over_ranks


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = synth.ragged_lengths(37, 32000, 240000, seed=9)
    shards = shard_utterances(lengths.tolist(), world)
    mine = shards[rank]
    # every rank derived the same partition without talking to anyone
    gathered = [None] * world
    dist.all_gather_object(gathered, shards)
    assert all(g == shards for g in gathered)
    units = float(sum(int(lengths[i]) for i in mine)) / 16000.0
    seconds = 0.010 * (rank + 1)               # rank 1 is the slow one
    tp = job_throughput(units, seconds)
    assert abs(max_over_ranks(seconds) - 0.020) < 1e-12
    assert abs(sum_over_ranks(units) - float(lengths.sum()) / 16000.0) < 1e-6
    if rank == 0:
        out.put((shards, tp, float(lengths.sum()) / 16000.0))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_aggregation():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    shards, tp, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    lengths = synth.ragged_lengths(37, 32000, 240000, seed=9)
    flat = sorted(i for s in shards for i in s)
    assert flat == list(range(37))                                   # a partition
    loads = [sum(int(lengths[i]) for i in s) for s in shards]
    assert abs(loads[0] - loads[1]) <= int(lengths.max())            # balanced to within one clip
    assert abs(tp - total / 0.020) < 1e-6                            # all units / slowest rank


def test_shard_single_rank_is_identity():
    assert shard_utterances([5, 3, 9], 1) == [[0, 1, 2]]
    assert shard_utterances([], 4) == [[], [], [], []]
