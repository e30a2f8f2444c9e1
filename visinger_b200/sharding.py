"""Host-side sharding of independent utterances across GPUs.

The hot path has no cross-utterance dependency (models/visinger.py:71-112; the reference itself runs batch 1,
tasks/visinger.py:246), so multi-GPU inference is "shard the utterance list, replicate the weights, gather the
waveforms on the host".  There is no collective on the data path; torch.distributed is only used for the barrier,
the max-over-ranks of device-measured times and (optionally) a host-side gather of results.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of utterance indices to ranks.

    Convolution cost is linear in the frame count, so balancing the summed length balances the work.  Deterministic:
    ties are broken by utterance index, then by rank.
    """
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(lengths[i])
    return shards


def bucket_by_length(indices: Sequence[int], lengths: Sequence[int], max_frames_per_batch: int,
                     max_batch: int = 64) -> List[List[int]]:
    """Length-sorted batches (the reference sorts by length too: utils/commons/dataset_utils.py:181-191) whose padded
    size `len(batch) * max(len)` stays under `max_frames_per_batch`."""
    order = sorted(indices, key=lambda i: (-int(lengths[i]), i))
    batches, cur = [], []
    for i in order:
        longest = int(lengths[cur[0]]) if cur else int(lengths[i])
        if cur and ((len(cur) + 1) * longest > max_frames_per_batch or len(cur) >= max_batch):
            batches.append(cur)
            cur = []
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches


def pad_batch(lengths: Sequence[int]) -> Tuple[int, List[int]]:
    """(padded length, valid lengths) of one batch -- right padding only, as the reference's collate_1d does."""
    return (max(int(n) for n in lengths) if len(lengths) else 0), [int(n) for n in lengths]


def max_over_ranks(value: float, device=None) -> float:
    """MAX-reduce a host float over the default process group (identity when not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_lengths(local: Sequence[int]) -> List[List[int]]:
    """All-gather small python lists (e.g. produced sample counts) for the host-side result bookkeeping."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [list(local)]
    out: List[List[int]] = [None] * dist.get_world_size()  # type: ignore[list-item]
    dist.all_gather_object(out, list(local))
    return out
