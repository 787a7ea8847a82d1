"""Multi-GPU partitioning of the inference path: utterances are independent units, so ranks take disjoint
(fs, length)-bucketed subsets and NO data-path collective exists.  Buckets follow the reference's
GroupedBatchSampler idea (dataset.py:338-401): one sample rate per batch, length-sorted, rank-strided."""
from __future__ import annotations

import torch


def shard_utterances(lengths, sample_rates, rank, world_size, max_batch=64, max_pad_ratio=0.0):
    """-> list of (fs, [utterance indices]) batches owned by `rank`: same fs per batch, lengths sorted descending,
    indices dealt round-robin over ranks (``sorted_indices[rank::world_size]``, reference dataset.py:361).

    max_pad_ratio bounds (longest - shortest) / longest inside a batch.  The default 0.0 batches only utterances of
    EQUAL length: the reference's inference loop is batch-1 (inference.py:48-64), and right-zero-padding a shorter
    utterance changes its output (the padding takes part in every GroupNorm statistic and the backward-direction
    BLSTMs run through it).  Equal-length batches reproduce the batch-1 result of every member; a
    positive ratio trades that for fuller batches (the training collate's semantics, dataset.py:404-441)."""
    by_fs = {}
    for i, (n, fs) in enumerate(zip(lengths, sample_rates)):
        by_fs.setdefault(int(fs), []).append((int(n), i))
    batches = []
    for fs in sorted(by_fs):
        order = sorted(by_fs[fs], key=lambda p: (-p[0], p[1]))
        mine = order[rank::world_size]
        cur, longest = [], 0
        for n, i in mine:
            if cur and (len(cur) >= max_batch or (longest - n) > max_pad_ratio * longest):
                batches.append((fs, cur))
                cur = []
            if not cur:
                longest = n
            cur.append(i)
        if cur:
            batches.append((fs, cur))
    return batches


def shard_batch(n_items, rank, world_size):
    """Contiguous split of one batch of `n_items` independent utterances over ranks (BASELINE config 2: 64 utterances,
    64/G per GPU) -> (first, count) of `rank`; the first n_items % world ranks take one extra."""
    base, rem = divmod(int(n_items), int(world_size))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


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
