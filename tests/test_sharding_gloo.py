"""CPU: the N>1 host logic (utterance sharding, bucketing, max-over-ranks) with a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from visinger_b200.sharding import shard_utterances, bucket_by_length, max_over_ranks, gather_lengths


def _lengths(n=512, seed=1234):
    rng = np.random.default_rng(seed)   # SURVEY.md 8(d) config 4/5 length distribution
    return np.clip(np.round(80 * rng.lognormal(np.log(5.6), 0.45, n)), 120, 1280).astype(int).tolist()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_shard_utterances_partitions_and_balances(world):
    L = _lengths()
    shards = shard_utterances(L, world)
    flat = sorted(i for s in shards for i in s)
    assert flat == list(range(len(L)))                       # a partition: every utterance exactly once
    loads = [sum(L[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(L)                 # LPT bound
    assert shard_utterances(L, world) == shards              # deterministic


def test_shard_edge_cases():
    assert shard_utterances([], 4) == [[], [], [], []]
    assert shard_utterances([5], 3) == [[0], [], []]
    with pytest.raises(ValueError):
        shard_utterances([1, 2], 0)


def test_bucketing_bounds_padding():
    L = _lengths(200)
    batches = bucket_by_length(range(len(L)), L, max_frames_per_batch=16000, max_batch=64)
    assert sorted(i for b in batches for i in b) == list(range(len(L)))
    for b in batches:
        assert len(b) * max(L[i] for i in b) <= 16000 or len(b) == 1
        assert len(b) <= 64
    padded = sum(len(b) * max(L[i] for i in b) for b in batches)
    assert padded <= 1.25 * sum(L)                           # length-sorted buckets waste < 25 % on padding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = _lengths(64)
        mine = shard_utterances(L, world)[rank]
        produced = [L[i] * 300 for i in mine]                 # samples this rank would synthesise
        everyone = gather_lengths(produced)
        t = max_over_ranks(10.0 + rank)                       # the slowest rank defines the step time
        dist.barrier()
        q.put((rank, sorted(mine), sum(sum(x) for x in everyone), t))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_run():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    L = _lengths(64)
    assert sorted(i for _, m, _, _ in res for i in m) == list(range(64))
    for _, _, total, t in res:
        assert total == sum(L) * 300                           # every rank sees the whole job's sample count
        assert t == 11.0                                       # MAX over ranks
