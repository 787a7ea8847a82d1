"""Multi-GPU partitioning of the inference path: utterances are independent units, so ranks take disjoint
(fs, length)-bucketed subsets and NO data-path collective exists.  Buckets follow the reference's
GroupedBatchSampler idea (dataset.py:338-401): one sample rate per batch, length-sorted, rank-strided."""
from __future__ import annotations

import torch


def shard_utterances(lengths, sample_rates, rank, world_size, max_batch=64):
    """-> list of (fs, [utterance indices]) batches owned by `rank`: same fs per batch, lengths sorted descending,
    indices dealt round-robin over ranks (``sorted_indices[rank::world_size]``, reference dataset.py:361)."""
    by_fs = {}
    for i, (n, fs) in enumerate(zip(lengths, sample_rates)):
        by_fs.setdefault(int(fs), []).append((int(n), i))
    batches = []
    for fs in sorted(by_fs):
        order = [i for _, i in sorted(by_fs[fs], key=lambda p: (-p[0], p[1]))]
        mine = order[rank::world_size]
        for s in range(0, len(mine), max_batch):
            batches.append((fs, mine[s:s + max_batch]))
    return batches


def gather_max_ms(ms: float) -> float:
    """Max over ranks of a per-rank device time (the only cross-rank exchange of the inference bench)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms)
    t = torch.tensor([ms], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
